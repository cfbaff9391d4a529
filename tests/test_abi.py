"""CPU checks of the drop-in boundary: the library loads, exports every declared symbol, ctypes mirrors match."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from posetraj_b200 import build, _lib
    build.build()
    return _lib.lib()


def declared_symbols():
    text = (ROOT / "include" / "posetraj_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pt_[a-z0-9_]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    from posetraj_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 6
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/posetraj_b200.h but not exported"
    assert sorted(_lib.exported_symbols()) == names


def test_struct_sizes_match(lib):
    from posetraj_b200 import _lib
    for name in dir(_lib):
        cls = getattr(_lib, name)
        if isinstance(cls, type) and issubclass(cls, ctypes.Structure) and name.startswith("Pt"):
            assert lib.pt_sizeof(name.encode()) == ctypes.sizeof(cls), name


def test_argument_errors_are_value_errors(lib):
    from posetraj_b200 import _lib
    with pytest.raises(ValueError):
        _lib.check(lib.pt_gemm(None, None), "pt_gemm")
    assert b"null" in lib.pt_last_error()


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from posetraj_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.PoseTrajLibError):
        _lib.lib()


def test_product_package_never_touches_the_oracle():
    """oracle/ is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it."""
    offenders = []
    for path in list((ROOT / "posetraj_b200").rglob("*.py")) + list((ROOT / "posetraj_b200").rglob("*.cu")) + \
            list((ROOT / "tools").rglob("*.py")):
        text = path.read_text()
        if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M):
            offenders.append(str(path.relative_to(ROOT)))
    assert offenders == []
    bench = (ROOT / "bench.py").read_text()
    # bench.py: the oracle is imported only inside the two BASELINE functions — the CPU baseline (also the --impl
    # reference arm) and the torch-eager-bf16 "library bar" the round-1 verdict asked for — never on the product path
    allowed = {"cpu_oracle_steps_per_sec", "library_baseline_leg"}
    for chunk in re.split(r"\n(?=def )", bench):
        name = re.match(r"def (\w+)", chunk)
        if re.search(r"^\s*(from|import)\s+oracle\b", chunk, flags=re.M):
            assert name is not None and name.group(1) in allowed, chunk[:80]
