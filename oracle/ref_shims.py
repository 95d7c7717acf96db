"""Import shims that let the UNMODIFIED reference (/root/reference) be imported in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (dynamicpdb_b200/) may import this file.
It is used by oracle/make_golden.py (fixture generation) and by the CPU tests that pin the oracle
restatement against the live reference when /root/reference is mounted.

The reference imports five packages at module-import time that are not installed here
(SURVEY.md §8c): tree (openfold/np/residue_constants.py:24), deepspeed
(openfold/model/primitives.py:21, openfold/utils/checkpointing.py:15), Bio
(openfold/np/protein.py:24), ml_collections (openfold/utils/loss.py:18), omegaconf
(src/data/utils.py:7,15).  None of them is used by the score-network arithmetic.
"""
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DFOLD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "model"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    m.__path__ = []  # behave like a package so "import a.b" works
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, dict):
        return {k: _map_structure(fn, *[s[k] for s in structs]) for k in s0}
    if isinstance(s0, (list, tuple)):
        out = [_map_structure(fn, *xs) for xs in zip(*structs)]
        return type(s0)(out) if not hasattr(s0, "_fields") else type(s0)(*out)
    return fn(*structs)


class _Anything:
    """Inert attribute sink for Bio / ml_collections / omegaconf names touched at import time."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


def _lenient_getattr(name):
    if name.startswith("__") and name.endswith("__"):
        raise AttributeError(name)
    return _Anything()


def install():
    """Install the stubs and put the reference on sys.path (after whatever is already there)."""
    import torch  # noqa: F401  (must be fully imported before inert stubs exist: torch inspects sys.modules)
    try:
        import tree  # noqa: F401
    except Exception:
        _stub("tree", map_structure=_map_structure)
    try:
        import deepspeed  # noqa: F401
    except Exception:
        ds = _stub("deepspeed")
        ds.utils = _stub("deepspeed.utils", is_initialized=lambda: False)
        ds.checkpointing = _stub("deepspeed.checkpointing", is_configured=lambda: False,
                                 checkpoint=lambda fn, *a: fn(*a))
        ds.comm = _stub("deepspeed.comm", is_initialized=lambda: False)
    for pkg, subs in {
        "Bio": ["Bio.PDB", "Bio.PDB.Chain", "Bio.PDB.PDBParser", "Bio.PDB.Polypeptide", "Bio.Data",
                "Bio.Data.SCOPData", "Bio.SVDSuperimposer"],
        "ml_collections": [],
        "omegaconf": [],
    }.items():
        try:
            __import__(pkg)
            continue
        except Exception:
            pass
        root = _stub(pkg)
        root.__getattr__ = _lenient_getattr  # type: ignore[attr-defined]
        for s in subs:
            m = _stub(s)
            m.__getattr__ = _lenient_getattr  # type: ignore[attr-defined]
    if reference_available() and REFERENCE_ROOT not in sys.path:
        sys.path.append(REFERENCE_ROOT)


def purge_reference_modules():
    """Drop cached `src.*` / `openfold.*` modules so a different overlay order can be imported."""
    for k in list(sys.modules):
        if k == "src" or k.startswith("src.") or k == "openfold" or k.startswith("openfold."):
            del sys.modules[k]
