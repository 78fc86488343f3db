"""Every native launcher the ops layer uses exists (guards against an accidentally truncated module)."""
import ast
import os

from mp_former_b200 import native, ops


def test_ops_only_reference_existing_native_functions():
    src = open(ops.__file__).read()
    used = {n.attr for n in ast.walk(ast.parse(src))
            if isinstance(n, ast.Attribute) and isinstance(n.value, ast.Name) and n.value.id == "native"}
    missing = sorted(u for u in used if not hasattr(native, u))
    assert not missing, missing
    assert os.path.exists(native.__file__)
